"""GPU tests of the block-centric kernels (csrc/blk_attn.cuh):
  * x2_blocks_build: the closed blocks of a molecule's line graph are "the bonds leaving atom j" + the bonds
    entering j, verified property by property against a host restatement; arbitrary edge_index either yields
    valid blocks or none;
  * the factorised lin_sbf (F_B_2D's output is table[src] * Y_l0(theta); lin_sbf is
    evaluated inside the attention kernels) against the dense [T, S] path and the fp64 oracle: 1e-5."""
import numpy as np
import pytest
import torch

from oracle import conv as oconv
from util import FP32_TOL, relerr

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _factorised_on():
    """The factorised lin_sbf is opt-in (sbftransformer_conv.USE_FACTORS / X2GNN_SGF=1): on for these tests."""
    import x2gnn_b200.sbftransformer_conv as sc
    old = sc.USE_FACTORS
    sc.USE_FACTORS = True
    yield
    sc.USE_FACTORS = old

DIMS = (128, 16, 42, 6, 128)


def _graph(nmol, seed, ball=False):
    from x2gnn_b200 import synth
    from x2gnn_b200.edge_graph import vertex_to_edge_2
    b = synth.ball_batch(nmol, n_atoms=120, seed=seed) if ball else synth.qm9_batch(nmol, seed=seed)
    ei = torch.from_numpy(b["edge_index"]).cuda()
    N = len(b["x"])
    tri, aj, ai, ak = vertex_to_edge_2(ei, N)
    return b, ei, N, tri, aj, ai, ak


def test_blocks_of_a_molecule_batch_are_the_atoms():
    from x2gnn_b200 import graph_meta
    b, ei, N, tri, aj, ai, ak = _graph(5, seed=3)
    E = ei.size(1)
    meta = graph_meta.build(tri, E, want_blocks=True)
    blk = meta.blocks
    assert blk is not None and meta.target_sorted
    sptr, tptr, tord = blk.sptr.cpu().numpy(), blk.tptr.cpu().numpy(), blk.tord.cpu().numpy()
    src, tgt = tri[0].cpu().numpy(), tri[1].cpu().numpy()
    nb = blk.n
    assert sptr[0] == 0 and sptr[nb] == E and tptr[0] == 0 and tptr[nb] == E
    assert np.all(np.diff(sptr[:nb + 1]) >= 1)
    blk_of_src = np.repeat(np.arange(nb), np.diff(sptr[:nb + 1]))
    blk_of_tgt = np.empty(E, dtype=np.int64)
    for k in range(nb):
        seg = tord[tptr[k]:tptr[k + 1]]
        assert np.all(np.diff(seg) > 0)                       # ascending target ids inside a block
        blk_of_tgt[seg] = k
    assert sorted(tord.tolist()) == list(range(E))            # every line-node is a target of exactly one block
    assert np.array_equal(blk_of_src[src], blk_of_tgt[tgt])   # closure: source and target of a triplet share a block
    # for a molecule the block of a source bond (j -> k) is its first atom j whenever deg(j) >= 3
    first = ei[0].cpu().numpy()
    deg = np.bincount(first, minlength=N)
    big = deg[first] >= 3
    for k in np.unique(blk_of_src[big]):
        atoms = np.unique(first[(blk_of_src == k)])
        assert len(atoms) == 1
    cnt = np.bincount(blk_of_src[src], minlength=nb)
    assert blk.max_triplets == cnt.max() and blk.max_src == np.diff(sptr[:nb + 1]).max()


def test_blocks_rejected_or_valid_for_arbitrary_graphs():
    from x2gnn_b200 import graph_meta
    g = torch.Generator().manual_seed(0)
    E, T = 300, 4000
    tgt = torch.sort(torch.randint(0, E, (T,), generator=g)).values
    src = torch.randint(0, E, (T,), generator=g)
    meta = graph_meta.build(torch.stack([src, tgt]).cuda(), E, want_blocks=True)
    if meta.blocks is not None:        # random sources: essentially never closed, but if so it must be valid
        blk = meta.blocks
        sp = blk.sptr.cpu().numpy()
        bs = np.repeat(np.arange(blk.n), np.diff(sp[:blk.n + 1]))
        bt = np.empty(E, dtype=np.int64)
        to, tp = blk.tord.cpu().numpy(), blk.tptr.cpu().numpy()
        for k in range(blk.n):
            bt[to[tp[k]:tp[k + 1]]] = k
        assert np.array_equal(bs[src.numpy()], bt[tgt.numpy()])
    # unsorted targets: no blocks
    meta2 = graph_meta.build(torch.stack([src, tgt.flip(0)]).cuda(), E, want_blocks=True)
    assert meta2.blocks is None


def _layer(state=None, seed=0, **kw):
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv
    D, H, S, R, A = DIMS
    torch.manual_seed(seed)
    c = SBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A, **kw)
    with torch.no_grad():
        for p in c.parameters():
            if p.dim() == 1:
                p.uniform_(-0.2, 0.2)
    if state is not None:
        c.load_state_dict(state)
    return c.cuda()


def _fb_inputs(nmol, seed, ball=False):
    """Inputs of one layer call the way xgnn.py:38-71 produces them: sbf = F_B_2D(d, angle, triplets[0])."""
    from x2gnn_b200.angular_basis_layer import F_B_2D
    b, ei, N, tri, aj, ai, ak = _graph(nmol, seed, ball)
    pos = torch.from_numpy(b["atom_pos"]).cuda()
    d = (pos[ei[0]] - pos[ei[1]]).norm(dim=1)
    ji, jk = pos[ai] - pos[aj], pos[ak] - pos[aj]
    ang = torch.atan2(torch.linalg.cross(ji, jk).norm(dim=1), (ji * jk).sum(1))
    sbf = F_B_2D(7, 6, 5.0)(d, ang, tri[0])
    E, T = ei.size(1), tri.size(1)
    g = torch.Generator(device="cuda").manual_seed(seed)
    D, H, S, R, A = DIMS
    x = torch.randn(E, D, device="cuda", generator=g)
    rbf = torch.rand(E, R, device="cuda", generator=g) * 2 - 1
    ea = torch.randn(T, A, device="cuda", generator=g)
    gout = torch.randn(E, D, device="cuda", generator=g)
    return dict(x=x, rbf=rbf, sbf=sbf, edge_attr=ea, edge_index=tri, grad_out=gout, atom_j=aj, N=N)


def _step(layer, rec, sbf, edge_attr=None, edge_attr_index=None):
    layer.zero_grad(set_to_none=True)
    x = rec["x"].clone().requires_grad_(True)
    rbf = rec["rbf"].clone().requires_grad_(True)
    ea = (rec["edge_attr"] if edge_attr is None else edge_attr).clone().requires_grad_(True)
    out = layer(sbf, rbf, x=x, edge_index=rec["edge_index"], edge_attr=ea, edge_attr_index=edge_attr_index)
    out.backward(rec["grad_out"])
    grads = {"x": x.grad, "rbf": rbf.grad, "edge_attr": ea.grad}
    grads.update({k: p.grad.clone() for k, p in layer.named_parameters()})
    return out.detach().clone(), grads


@pytest.mark.parametrize("table", [False, True])
@pytest.mark.parametrize("ball", [False, True])
def test_factorised_sbf_matches_dense_and_oracle(table, ball):
    import x2gnn_b200.sbftransformer_conv as sc
    rec = _fb_inputs(2 if ball else 6, seed=2, ball=ball)
    layer = _layer(seed=1)
    kw = {}
    if table:          # segment-constant edge_attr: one row per atom, indexed by the target bond's central atom
        g = torch.Generator(device="cuda").manual_seed(5)
        tab = torch.randn(rec["N"], DIMS[4], device="cuda", generator=g)
        E = rec["x"].size(0)
        idx = torch.zeros(E, dtype=torch.int64, device="cuda")
        idx[rec["edge_index"][1]] = rec["atom_j"]
        kw = dict(edge_attr=tab, edge_attr_index=idx)
    sbf = rec["sbf"]
    assert getattr(sbf, "_x2_factors", None) is not None
    before = sc.PLAN_COUNTS["factorised"]
    o_f, g_f = _step(layer, rec, sbf, **kw)
    assert sc.PLAN_COUNTS["factorised"] == before + 1          # the factorised kernels really ran
    o_d, g_d = _step(layer, rec, sbf.clone(), **kw)            # the clone carries no factors: dense [T, S] path
    assert sc.PLAN_COUNTS["factorised"] == before + 1
    assert relerr(o_f, o_d) < FP32_TOL
    for k in g_d:
        if k == "lin_key.bias":
            continue
        assert relerr(g_f[k], g_d[k]) < 2 * FP32_TOL, k       # two fp32 paths, each within 1e-5 of the fp64 oracle (below)
    # twice the same bits
    o_f2, g_f2 = _step(layer, rec, sbf, **kw)
    assert torch.equal(o_f, o_f2)
    for k in g_f:
        assert torch.equal(g_f[k], g_f2[k]), k
    # fp64 oracle on the dense tensor
    D, H, S, R, A = DIMS
    ref = oconv.OracleSBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A).double()
    ref.load_state_dict({k: v.double().cpu() for k, v in layer.state_dict().items()})
    x = rec["x"].double().cpu().requires_grad_(True)
    rbf = rec["rbf"].double().cpu().requires_grad_(True)
    if table:
        ea = kw["edge_attr"].double().cpu().requires_grad_(True)
        ea_t = ea[kw["edge_attr_index"].cpu()[rec["edge_index"][1].cpu()]]
    else:
        ea = rec["edge_attr"].double().cpu().requires_grad_(True)
        ea_t = ea
    o_r = ref(sbf.detach().double().cpu(), rbf, x=x, edge_index=rec["edge_index"].cpu(), edge_attr=ea_t)
    o_r.backward(rec["grad_out"].double().cpu())
    assert relerr(o_f, o_r) < FP32_TOL
    assert relerr(g_f["x"], x.grad) < FP32_TOL
    assert relerr(g_f["rbf"], rbf.grad) < FP32_TOL
    assert relerr(g_f["edge_attr"], ea.grad) < FP32_TOL
    for k, p in ref.named_parameters():
        if k != "lin_key.bias":
            assert relerr(g_f[k], p.grad) < FP32_TOL, k
            # the dense path on the same graph (2e5 triplets on the ball graphs: the T-row weight gradients were
            # 1.3e-5 off before the tensor-core accumulation was cut into periods, csrc/tc_gemm.cuh kG2Period)
            assert relerr(g_d[k], p.grad) < FP32_TOL, k


def test_factors_are_dropped_when_they_do_not_apply():
    import x2gnn_b200.sbftransformer_conv as sc
    rec = _fb_inputs(3, seed=4)
    layer = _layer(seed=2)
    sbf = rec["sbf"]
    before = sc.PLAN_COUNTS["factorised"]
    # (1) sbf modified in place after F_B_2D: the factors no longer describe it
    sbf2 = rec["sbf"].clone()
    sbf2._x2_factors = sbf._x2_factors
    sbf2.mul_(2.0)
    o_a, _ = _step(layer, rec, sbf2)
    o_b, _ = _step(layer, rec, (rec["sbf"] * 2.0))
    assert sc.PLAN_COUNTS["factorised"] == before
    assert torch.equal(o_a, o_b)
    # (2) a different edge_index (same shape, sources permuted inside nothing): value check fails -> dense
    tri = rec["edge_index"].clone()
    rec2 = dict(rec, edge_index=tri)            # equal values, different storage: accepted after one comparison
    o_c, _ = _step(layer, rec2, sbf)
    assert sc.PLAN_COUNTS["factorised"] == before + 1
    o_d, _ = _step(layer, rec, sbf.clone())
    assert relerr(o_c, o_d) < FP32_TOL
    # (3) attention weights requested: dense path
    out, (ei_, alpha) = layer(sbf, rec["rbf"], x=rec["x"], edge_index=rec["edge_index"], edge_attr=rec["edge_attr"],
                              return_attention_weights=True)
    assert alpha.shape == (rec["edge_index"].size(1), DIMS[1])
    assert sc.PLAN_COUNTS["factorised"] == before + 1


def _random_blocked_graph(seed, nblocks=40, with_gaps=True):
    """A target-sorted line graph made of closed blocks that are NOT molecule-shaped: block b owns a contiguous run of
    source nodes and an arbitrary set of target nodes; every target's segment is a sorted subset of its block's
    sources (with consecutive sources present somewhere, so that the block hangs together), some targets and some
    sources are unused."""
    rng = np.random.default_rng(seed)
    E = 0
    blocks = []
    for _ in range(nblocks):
        ns = int(rng.integers(1, 12))
        blocks.append((E, ns))
        E += ns
    extra = int(rng.integers(0, 5))            # line-nodes that are never a source of anything
    E += extra
    tgt_of_block = np.array_split(rng.permutation(E), nblocks)      # every node is a target of some block
    src, tgt = [], []
    for (s0, ns), tg in zip(blocks, tgt_of_block):
        first = True
        for e in tg:
            if with_gaps and rng.random() < 0.15 and not first:
                continue                                            # empty segment
            if first:
                seg = np.arange(ns)                                 # one full segment links the block's sources
                first = False
            else:
                seg = np.flatnonzero(rng.random(ns) < 0.6)
            for k in seg:
                src.append(s0 + k)
                tgt.append(int(e))
    src, tgt = np.array(src), np.array(tgt)
    order = np.lexsort((src, tgt))
    return torch.from_numpy(np.stack([src[order], tgt[order]])).long(), E


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_factorised_sbf_on_arbitrary_closed_blocks(seed):
    """Nothing in the block kernels assumes a molecule: random closed blocks (segments = arbitrary subsets of a block's
    sources, empty segments, unused sources, nT != nS), factors supplied by hand, against the dense path."""
    import x2gnn_b200.sbftransformer_conv as sc
    from x2gnn_b200 import graph_meta
    from x2gnn_b200.angular_basis_layer import SbfFactors
    ei, E = _random_blocked_graph(seed)
    ei = ei.cuda()
    T = ei.size(1)
    meta = graph_meta.build(ei, E, want_blocks=True)
    assert meta.blocks is not None and meta.target_sorted
    D, H, S, R, A = DIMS
    g = torch.Generator(device="cuda").manual_seed(seed)
    table = torch.randn(E, S, device="cuda", generator=g)
    ang = torch.rand(T, device="cuda", generator=g) * 3.1
    from x2gnn_b200 import _lib
    sbf = torch.empty(T, S, device="cuda")
    _lib.check(_lib.lib().x2_sbf_fwd(_lib.ptr(table), _lib.ptr(ang), _lib.ptr(ei[0].contiguous()), T, E, 7, 6,
                                     _lib.ptr(sbf), _lib.stream()), "x2_sbf_fwd")
    sbf._x2_factors = SbfFactors(table, ang, ei[0], 7, 6, sbf._version)
    rec = dict(x=torch.randn(E, D, device="cuda", generator=g), rbf=torch.rand(E, R, device="cuda", generator=g),
               edge_attr=torch.randn(T, A, device="cuda", generator=g), edge_index=ei,
               grad_out=torch.randn(E, D, device="cuda", generator=g))
    layer = _layer(seed=seed)
    before = sc.PLAN_COUNTS["factorised"]
    o_f, g_f = _step(layer, rec, sbf)
    assert sc.PLAN_COUNTS["factorised"] == before + 1
    o_d, g_d = _step(layer, rec, sbf.clone())
    assert relerr(o_f, o_d) < FP32_TOL
    for k in g_d:
        if k != "lin_key.bias":
            assert relerr(g_f[k], g_d[k]) < 2 * FP32_TOL, k
