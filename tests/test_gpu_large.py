"""GPU parity at the larger BASELINE.json sizes, through size-independent properties plus an oracle
comparison on what the oracle can do in seconds:
  configs[2]  OCELOT-sized molecules (~100 atoms), triplet edge-graph aggregation
  configs[3]  500-atom ball-packed graphs, SBF-conv layer
  configs[1]  the full QM9 batch-128 graph (bench workload): determinism + linearity in G
"""
import numpy as np
import pytest
import torch

from oracle import conv as oconv, graph as ograph
from util import FP32_TOL, relerr

pytestmark = pytest.mark.gpu


def _layer(dims=(128, 16, 42, 6, 128), seed=0):
    from x2gnn_b200.sbftransformer_conv import SBFTransformerConv
    D, H, S, R, A = dims
    torch.manual_seed(seed)
    ref = oconv.OracleSBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A)
    mine = SBFTransformerConv(D, D // H, heads=H, sbf_dim=S, rbf_dim=R, edge_dim=A)
    mine.load_state_dict(ref.state_dict())
    return ref, mine.cuda()


def _inputs(E, tri, dims=(128, 16, 42, 6, 128), seed=0):
    from x2gnn_b200 import synth
    D, H, S, R, A = dims
    ci = synth.conv_inputs(E, tri, D, S, R, A, seed=seed)
    return {k: torch.from_numpy(v) for k, v in ci.items()}


def test_ocelot_sized_molecules_triplets_and_conv():
    """~100-atom molecules: long segments (up to ~90 triplets per target)."""
    from x2gnn_b200 import edge_graph, synth
    rng = np.random.default_rng(7)
    mols = [synth.synth_mol(int(n), rng) for n in (84, 100, 120)]
    b = synth.collate(mols, seed=7)
    ei = torch.from_numpy(b["edge_index"])
    N, E = len(b["x"]), ei.size(1)
    got = edge_graph.vertex_to_edge_2(ei.cuda(), N)
    want = ograph.vertex_to_edge_2(ei, N)
    for a, w in zip(got, want):
        assert torch.equal(a.cpu(), w)                       # bit-exact
    tri = want[0]
    seg = torch.bincount(tri[1], minlength=E)
    assert int(seg.max()) > 40
    ref, mine = _layer()
    inp = _inputs(E, tri.numpy(), seed=1)
    o_ref = ref.double()(inp["sbf"].double(), inp["rbf"].double(), x=inp["x"].double(), edge_index=tri,
                         edge_attr=inp["edge_attr"].double())
    with torch.no_grad():
        o = mine(inp["sbf"].cuda(), inp["rbf"].cuda(), x=inp["x"].cuda(), edge_index=tri.cuda(),
                 edge_attr=inp["edge_attr"].cuda())
    assert relerr(o, o_ref) < FP32_TOL


def test_ball500_graph_properties_and_conv_forward():
    """BASELINE configs[3]: one 500-atom graph (E ~ 19k, T ~ 0.75M)."""
    from x2gnn_b200 import atom_graph, edge_graph, synth
    b = synth.ball_batch(1, n_atoms=500, seed=0)
    pos = torch.from_numpy(b["atom_pos"]).cuda()
    ei, _ = atom_graph.radius_graph(pos, None, 5.0)
    assert torch.equal(ei.cpu(), torch.from_numpy(b["edge_index"]))
    tri, aj, ai, ak = edge_graph.vertex_to_edge_2(ei, 500)
    E, T = ei.size(1), tri.size(1)
    # T = sum over bonds (i->j) of deg(j) - 1 for a symmetric radius graph
    deg = torch.bincount(ei[0], minlength=500)
    assert T == int((deg[ei[1]] - 1).sum())
    assert bool((tri[1, 1:] >= tri[1, :-1]).all())           # target-sorted
    assert bool((ei[1][tri[1]] == ei[0][tri[0]]).all())      # dst(e) == src(f) == j
    assert bool((aj == ei[1][tri[1]]).all() and (ai == ei[0][tri[1]]).all() and (ak == ei[1][tri[0]]).all())
    assert bool((ai != ak).all())
    # conv forward vs oracle (fp32 CPU oracle, a few seconds)
    ref, mine = _layer()
    inp = _inputs(E, tri.cpu().numpy(), seed=2)
    with torch.no_grad():
        o_ref = ref(inp["sbf"], inp["rbf"], x=inp["x"], edge_index=tri.cpu(), edge_attr=inp["edge_attr"])
        o = mine(inp["sbf"].cuda(), inp["rbf"].cuda(), x=inp["x"].cuda(), edge_index=tri, edge_attr=inp["edge_attr"].cuda())
    assert relerr(o, o_ref) < 3e-5                            # both sides fp32 here


def test_full_bench_workload_properties():
    """BASELINE configs[1] at full size: determinism and linearity of the backward in grad_out."""
    from x2gnn_b200 import synth
    b = synth.qm9_batch(128, seed=0)
    tri = torch.from_numpy(synth.triplets_host(b["edge_index"], len(b["x"]))[0])
    E = b["edge_index"].shape[1]
    _, mine = _layer()
    inp = {k: v.cuda() for k, v in _inputs(E, tri.numpy(), seed=0).items()}
    x = inp["x"].requires_grad_(True)
    ea = inp["edge_attr"].requires_grad_(True)
    params = list(mine.parameters())

    def run(g):
        out = mine(inp["sbf"], inp["rbf"], x=x, edge_index=inp["edge_index"], edge_attr=ea)
        return out, torch.autograd.grad(out, [x, ea] + params, g)

    gen = torch.Generator("cuda").manual_seed(0)
    g1 = torch.randn(E, 128, device="cuda", generator=gen)
    g2 = torch.randn(E, 128, device="cuda", generator=gen)
    o1, r1 = run(g1)
    o1b, r1b = run(g1)
    assert torch.equal(o1, o1b) and all(torch.equal(a, b_) for a, b_ in zip(r1, r1b))   # bitwise deterministic
    _, r2 = run(g2)
    _, r12 = run(g1 + 2 * g2)
    names = ["x", "edge_attr"] + [n for n, _ in mine.named_parameters()]
    for name, a, b_, c in zip(names, r1, r2, r12):             # backward is linear in grad_out
        if name == "lin_key.bias":                             # analytically zero (App. A): rounding noise only
            continue
        assert relerr(c, a.double() + 2 * b_.double()) < 2e-5, name
    # rows of targets without incoming triplets equal lin_skip(x)
    cnt = torch.bincount(inp["edge_index"][1], minlength=E)
    empty = (cnt == 0).nonzero().flatten()
    if empty.numel():
        want = torch.nn.functional.linear(x[empty].double(), mine.lin_skip.weight.double(), mine.lin_skip.bias.double())
        assert relerr(o1[empty], want) < FP32_TOL
