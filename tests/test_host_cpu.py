"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the header declares,
the host modules keep the reference's constructor / state_dict contract, and the product path
fails loudly (no CPU fallback, no oracle import)."""
import ast
import copy
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "x2-gnn_b200")


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "x2gnn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(x2_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    from x2gnn_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build with: python x2-gnn_b200/build.py"
    h = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(h, s), f"{s} declared in include/x2gnn.h but not exported"
    assert sorted(_lib.SIGNATURES) == syms          # ctypes table mirrors the header one to one
    assert _lib.lib().x2_version() >= 100


def test_workspace_queries_need_no_gpu():
    from x2gnn_b200 import _lib
    L = _lib.lib()
    assert L.x2_scan_workspace_bytes(1000) > 0
    assert L.x2_triplets_workspace_bytes(1000, 100) > 4 * 1000 * 4
    assert L.x2_meta_workspace_bytes(5000, 1000) > 4 * 1000 * 4
    d = _lib.ConvDesc()
    d.E, d.T, d.D, d.H, d.C, d.S, d.R, d.A = 100, 2000, 128, 16, 8, 42, 6, 128
    fwd = L.x2_sbfconv_fwd_workspace_bytes(ctypes.byref(d))
    bwd = L.x2_sbfconv_bwd_workspace_bytes(ctypes.byref(d))
    assert fwd >= 100 * 128 * 4 and bwd >= 2 * 2000 * 128 * 4


def test_conv_constructor_contract():
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv
    c = SBFTransformerConv(128, 8, heads=16, sbf_dim=42, rbf_dim=6, dropout=0, edge_dim=128)
    sd = c.state_dict()
    want = [("lin_key.weight", (128, 128)), ("lin_key.bias", (128,)), ("lin_query.weight", (128, 128)),
            ("lin_query.bias", (128,)), ("lin_value.weight", (128, 128)), ("lin_value.bias", (128,)),
            ("lin_edge.weight", (128, 128)), ("lin_skip.weight", (128, 128)), ("lin_skip.bias", (128,)),
            ("lin_sbf.weight", (128, 42)), ("lin_sbf.bias", (128,)), ("lin_rbf.weight", (128, 6))]
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == want
    assert sum(v.numel() for v in sd.values()) == 88704
    assert float(c.lin_sbf.bias.abs().max()) == 0.0
    assert repr(c) == "SBFTransformerConv(128, 8, heads=16)"
    c2 = copy.deepcopy(c)                                                   # EMA deep-copies the model
    assert all(torch.equal(a, b) for a, b in zip(c.state_dict().values(), c2.state_dict().values()))
    c.reset_parameters()
    for kw, absent in ((dict(edge_dim=None), "lin_edge.weight"), (dict(edge_dim=4, bias=False), "lin_skip.bias")):
        assert absent not in SBFTransformerConv(32, 8, heads=4, sbf_dim=6, rbf_dim=4, **kw).state_dict()
    b = SBFTransformerConv(32, 8, heads=4, sbf_dim=6, rbf_dim=4, edge_dim=4, beta=True, concat=False)
    assert tuple(b.state_dict()["lin_beta.weight"].shape) == (1, 24)
    assert tuple(b.state_dict()["lin_skip.weight"].shape) == (8, 32)


def test_conv_state_dict_matches_oracle_and_golden(golden):
    from oracle.conv import OracleSBFTransformerConv
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv
    rec = golden("conv")["cfg"]
    D, H, S, R, A = rec["dims"]
    mine = SBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A)
    mine.load_state_dict(rec["state_dict"], strict=True)     # a reference checkpoint loads unchanged
    ref = OracleSBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A)
    assert list(ref.state_dict()) == list(mine.state_dict())


def test_basis_modules_contract():
    from x2gnn_b200 import angular_basis_layer, basis_func, envelop, radial_basis_layer
    from oracle import bases as ob
    assert list(radial_basis_layer.RadialBasis(6, 5.0).state_dict()) == ["frequencies"]
    assert not list(radial_basis_layer.RadialBasis(6, 5.0, Trainable=False).state_dict())
    assert not list(angular_basis_layer.F_B_2D(7, 6, 5.0, 5).state_dict())
    e = envelop.poly_envelop(5.0, 5)
    assert (e.p, e.a, e.b, e.c) == (6, -28.0, 48, -21.0)
    z, n = basis_func.bessel_tables(7, 6)
    zo, no = ob.bessel_tables(7, 6)
    assert (z == zo).all() and (n == no.astype("float32")).all()
    assert abs(float(z[6, 5]) - 27.507868) < 1e-5 and abs(float(n[0, 0]) - 4.442883) < 1e-5   # App. B
    assert abs(basis_func.sph_harm_prefactor(2, 0) - 0.63078313) < 1e-7
    with pytest.raises(AssertionError):
        angular_basis_layer.F_B_2D(7, 65, 5.0)


def test_no_cpu_fallback():
    from x2gnn_b200 import _lib, edge_graph, envelop
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv
    if torch.cuda.is_available():
        pytest.skip("CPU-only behaviour")
    with pytest.raises(_lib.X2Error):
        edge_graph.vertex_to_edge_2(torch.zeros(2, 3, dtype=torch.long), 3)
    with pytest.raises(_lib.X2Error):
        envelop.poly_envelop(5.0, 5)(torch.ones(4))
    c = SBFTransformerConv(32, 8, heads=4, sbf_dim=6, rbf_dim=4, edge_dim=4)
    with pytest.raises(_lib.X2Error):
        c(torch.zeros(2, 6), torch.zeros(3, 4), x=torch.zeros(3, 32),
          edge_index=torch.zeros(2, 2, dtype=torch.long), edge_attr=torch.zeros(2, 4))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under the package (or the drop-in modules) may
    import it, nor read /root/reference."""
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if not f.endswith(".py"):
                continue
            src = open(os.path.join(dirpath, f)).read()
            assert "/root/reference" not in src, f
            for node in ast.walk(ast.parse(src)):
                names = []
                if isinstance(node, ast.Import):
                    names = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    names = [node.module or ""]
                assert not any(n == "oracle" or n.startswith("oracle.") for n in names), f


def test_dropin_install():
    import importlib
    import sys
    import x2gnn_b200
    x2gnn_b200.install()
    for m in x2gnn_b200.DROPIN_MODULES:
        mod = importlib.import_module(m)
        assert mod.__file__.startswith(x2gnn_b200.DROPIN_DIR)
    import sbftransformer_conv
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv
    assert sbftransformer_conv.SBFTransformerConv is SBFTransformerConv
    sys.path.remove(x2gnn_b200.DROPIN_DIR)
    for m in x2gnn_b200.DROPIN_MODULES:
        sys.modules.pop(m, None)


def test_sync_free_embedding_matches_nn_embedding():
    """EmbeddingBlock's graph-capturable lookup (distinct ids + histogram from XGNNPoly.prepare) against the
    stock nn.Embedding(max_norm, scale_grad_by_freq, padding_idx) it replaces: same renormalised table,
    same output, same gradients (atom_embedding.py of the reference; pure PyTorch, runs on the CPU)."""
    from x2gnn_b200.xgnn_model import EmbeddingBlock, graph_layer_norm
    torch.manual_seed(0)
    a = EmbeddingBlock(32)
    with torch.no_grad():
        a.embedding.weight.mul_(3.0)              # some rows above max_norm = 3, some below
        a.embedding.weight[6].mul_(0.05)
    b = copy.deepcopy(a)
    z = torch.tensor([1, 6, 6, 8, 1, 1, 7, 0, 6, 1])
    g = torch.randn(z.numel(), 32)
    ya = a(z)
    (ya * g).sum().backward()
    yb = b(z, torch.unique(z), torch.bincount(z, minlength=10))
    (yb * g).sum().backward()
    assert torch.allclose(a.embedding.weight, b.embedding.weight, rtol=1e-6, atol=0)   # in-place max_norm renorm
    assert float(a.embedding.weight[1].detach().norm()) < 3.001 < 4 < float(a.embedding.weight[9].detach().norm())   # row 9 is not in z
    assert torch.allclose(ya, yb, rtol=1e-5, atol=1e-6)
    assert torch.allclose(a.embedding.weight.grad, b.embedding.weight.grad, rtol=1e-5, atol=1e-6)
    assert torch.allclose(a.lin.weight.grad, b.lin.weight.grad, rtol=1e-5, atol=1e-6)
    x = torch.randn(7, 4)
    batch = torch.tensor([0, 0, 1, 1, 1, 3, 3])
    assert torch.equal(graph_layer_norm(x, batch, 4), graph_layer_norm(x, batch, 4, counts=torch.tensor([2, 3, 0, 2])))


def test_post_conv_entry_points_validate_and_refuse_cpu():
    """x2_graph_layernorm_* / x2_rbf_readout_*: argument errors are reported through the return code and
    x2_last_error before anything is launched (so they can be checked without a GPU); the Python bindings
    have no CPU path."""
    from x2gnn_b200 import _lib, graph_norm, readout_sum
    L = _lib.lib()
    assert L.x2_graph_layernorm_fwd(None, None, 1, 6, 1e-8, None, None, None) != 0      # D % 4 != 0
    assert b"multiple of 4" in L.x2_last_error()
    assert L.x2_graph_layernorm_fwd(None, None, 0, 128, 1e-8, None, None, None) == 0    # no graphs: nothing to do
    assert L.x2_graph_layernorm_bwd(None, None, None, 2, 128, None, None, None) != 0    # null pointers
    assert L.x2_rbf_readout_fwd(None, None, None, None, None, 4, 9, 64, 6, None, None) != 0   # D not 128 / 256
    assert b"D in {128, 256}" in L.x2_last_error()
    assert L.x2_rbf_readout_fwd(None, None, None, None, None, 4, 9, 128, 17, None, None) != 0  # R > 16
    assert L.x2_rbf_readout_bwd_workspace_bytes(2367, 43048, 128, 6) >= 128 * 7 * 4
    assert (L.x2_rbf_readout_bwd_workspace_bytes(2367, 43048, 256, 6)
            > L.x2_rbf_readout_bwd_workspace_bytes(2367, 43048, 128, 6) + 2 * 43048 * 6 * 4 - 1)
    assert readout_sum.supported(128, 6) and readout_sum.supported(256, 16)
    assert not readout_sum.supported(64, 6) and not readout_sum.supported(128, 17)
    rp = graph_norm.rowptr_from_counts(torch.tensor([3, 0, 2]))
    assert rp.dtype == torch.int32 and rp.tolist() == [0, 3, 3, 5]
    if torch.cuda.is_available():
        return
    with pytest.raises(_lib.X2Error):
        graph_norm.graph_layer_norm_rows(torch.zeros(5, 8), rp)
    with pytest.raises(_lib.X2Error):
        readout_sum.rbf_readout(torch.zeros(5, 128), torch.zeros(5, 6), torch.zeros(128, 6), None, rp)


def test_padded_width_embedding_is_exact_on_the_oracle():
    """SBFTransformerConv.padded_width: a layer whose in_channels / heads*out_channels the kernels are not
    instantiated for (reference ctor, sbftransformer_conv.py:19,47-48) is run zero-padded at the next kernel width.
    Here the SAME padding is applied to the CPU oracle in fp64: outputs, attention weights and every gradient of the
    padded layer, sliced, must equal the unpadded layer's (that is the exactness claim; the GPU test then checks the
    kernels at the padded width against the unpadded oracle)."""
    import torch.nn.functional as F
    from oracle import conv as oconv
    from x2gnn_b200 import sbftransformer_conv as sc

    assert sc.padded_width(128, 16, 8) == 128 and sc.padded_width(256, 16, 16) == 256
    assert sc.padded_width(48, 2, 16) == 64 and sc.padded_width(16, 1, 16) == 32
    assert sc.padded_width(200, 16, 8) == 256 and sc.padded_width(24, 3, 8) == 32
    assert sc.padded_width(300, 16, 8) is None and sc.padded_width(24, 1, 24) is None   # too wide / C not 2^k

    torch.manual_seed(0)
    in_ch, H, C, S, R, A = 24, 3, 8, 5, 3, 7
    Dp = sc.padded_width(in_ch, H, C)
    Hk = Dp // C
    ref = oconv.OracleSBFTransformerConv(in_ch, C, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A).double()
    p = {k: v.detach().clone().requires_grad_(True) for k, v in ref.state_dict().items()}
    E, T = 9, 40
    g = torch.Generator().manual_seed(1)
    ei = torch.stack([torch.randint(0, E, (T,), generator=g), torch.randint(0, E, (T,), generator=g).sort().values])
    x, rbf, sbf, ea = (torch.randn(n, w, generator=g, dtype=torch.float64).requires_grad_(True)
                       for n, w in ((E, in_ch), (E, R), (T, S), (T, A)))
    gout = torch.randn(E, H * C, generator=g, dtype=torch.float64)

    o0, a0 = oconv.sbfconv_forward(p, sbf, rbf, x, ei, ea, heads=H, out_channels=C)
    leaves = [x, rbf, sbf, ea] + list(p.values())
    g0 = torch.autograd.grad(o0, leaves, gout, allow_unused=True)

    pp = {}
    for k, v in p.items():
        if k == "lin_rbf.weight":
            pp[k] = sc._pad2(v, Dp, v.size(1))
        elif k in ("lin_edge.weight", "lin_sbf.weight"):
            pp[k] = sc._pad2(v, Dp, v.size(1))
        elif v.dim() == 2:
            pp[k] = sc._pad2(v, Dp, Dp)
        else:
            pp[k] = sc._pad1(v, Dp)
    o1, a1 = oconv.sbfconv_forward(pp, sbf, rbf, F.pad(x, (0, Dp - in_ch)), ei, ea, heads=Hk, out_channels=C)
    o1, a1 = o1[:, :H * C], a1[:, :H]
    g1 = torch.autograd.grad(o1, leaves, gout, allow_unused=True)
    assert torch.allclose(o0, o1, rtol=0, atol=1e-13) and torch.allclose(a0, a1, rtol=0, atol=1e-14)
    for u, v in zip(g0, g1):
        assert (u is None) == (v is None)
        if u is not None:
            assert torch.allclose(u, v, rtol=0, atol=1e-12)


def test_deferred_wgrad_bookkeeping_without_a_gpu():
    """tc_linear.WgradSlots never travels with a copied / pickled module (the EMA deepcopy of train_ema.py:45-47 must
    not drag the optimizer's buffers along); DeferredWgrads groups the 128-blocks of a layer by (rows, block width,
    alignment) and refuses a weight recorded twice before the flush."""
    import pickle
    from x2gnn_b200.tc_linear import DeferredWgrads, TCLinear, WgradSlots
    lin = TCLinear(338, 256)
    lin._x2_slots = WgradSlots(object(), torch.zeros(256, 338), torch.zeros(256))
    c = copy.deepcopy(lin)
    assert c._x2_slots is None and list(c.state_dict().keys()) == ["weight", "bias"]
    assert pickle.loads(pickle.dumps(lin._x2_slots)) is None
    q = DeferredWgrads()
    gy, x = torch.zeros(10, 256), torch.zeros(10, 338)
    gw, gb = torch.zeros(256, 338), torch.zeros(256)
    q.add(gy, x, 10, 256, 338, gw, gb)
    blocks = sorted((k[1], len(v)) for k, v in q.jobs.items())
    assert sum(n for _, n in blocks) == 6 and {kb for kb, _ in blocks} == {82, 128}     # 2 N blocks x K blocks 128|128|82
    jobs = [j for v in q.jobs.values() for j in v]
    assert sum(1 for j in jobs if j[6]) == 2                      # the bias gradient rides with the first K block only
    with pytest.raises(RuntimeError):
        q.add(gy, x, 10, 256, 338, gw, gb)
