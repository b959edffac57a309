"""GPU parity: integer kernels (radius graph, triplets, line-graph CSR) -- BIT-EXACT against the
oracle and the reference-generated golden vectors.  Calls go through the C-ABI library."""
import numpy as np
import pytest
import torch

from oracle import graph as ograph

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from x2gnn_b200 import atom_graph, edge_graph, graph_meta, synth
    return atom_graph, edge_graph, graph_meta, synth


@pytest.mark.parametrize("case", ["kat5", "mol9", "mol17", "mol29", "ball30"])
def test_golden_radius_and_triplets(golden, mods, case):
    atom_graph, edge_graph, _, _ = mods
    g = golden("graph")[case]
    D = atom_graph.calculate_Dij(g["pos"])
    assert D.is_cuda and D.dtype == torch.float32
    if "Dij" in g:   # distances: fp32 Gram arithmetic, a few ulp (BLAS-order dependent in the reference)
        ref = g["Dij"]
        ok = torch.isfinite(ref)
        assert torch.allclose(D.cpu()[ok], ref[ok], rtol=0, atol=5e-4)
    ei = atom_graph.gen_bonds_mini(D, g["cutoff"])
    assert ei.dtype == torch.int64
    assert torch.equal(ei.cpu(), g["edge_index"])
    # from the reference's own Dij as well
    if "Dij" in g:
        assert torch.equal(atom_graph.gen_bonds_mini(g["Dij"], g["cutoff"]).cpu(), g["edge_index"])
    tri, ej, ei_, ek = edge_graph.vertex_to_edge_2(g["edge_index"], g["pos"].size(0))
    for got, key in ((tri, "triplets_index"), (ej, "edge_j"), (ei_, "edge_i"), (ek, "edge_k")):
        assert got.dtype == torch.int64 and got.is_cuda
        assert torch.equal(got.cpu(), g[key].long()), key


@pytest.mark.parametrize("case", ["batch3", "directed_unsorted"])
def test_golden_triplets_batched_and_directed(golden, mods, case):
    _, edge_graph, _, _ = mods
    g = golden("graph")[case]
    tri, ej, ei_, ek = edge_graph.vertex_to_edge_2(g["edge_index"], g["num_nodes"])
    assert torch.equal(tri.cpu(), g["triplets_index"].long())
    assert torch.equal(ej.cpu(), g["edge_j"].long())
    assert torch.equal(ei_.cpu(), g["edge_i"].long())
    assert torch.equal(ek.cpu(), g["edge_k"].long())


def test_triplets_edge_cases(mods):
    _, edge_graph, _, _ = mods
    z = torch.zeros((2, 0), dtype=torch.int64)
    tri, ej, ei_, ek = edge_graph.vertex_to_edge_2(z, 4)          # no bonds
    assert tri.shape == (2, 0) and ej.numel() == 0
    dia = torch.tensor([[0, 1], [1, 0]])                           # diatomic: empty segments
    tri, *_ = edge_graph.vertex_to_edge_2(dia, 2)
    assert tri.shape == (2, 0)
    one = torch.tensor([[0, 1], [1, 2]])                           # exactly one triplet (ref squeezes to 0-d)
    tri, ej, ei_, ek = edge_graph.vertex_to_edge_2(one, 3)
    assert tri.tolist() == [[1], [0]] and (ej.tolist(), ei_.tolist(), ek.tolist()) == ([1], [0], [2])
    with pytest.raises(IndexError):
        edge_graph.vertex_to_edge_2(torch.tensor([[0, 5], [1, 0]]), 3)


@pytest.mark.parametrize("nmol,seed", [(8, 0), (128, 0)])
def test_triplets_qm9_batches_vs_oracle(mods, nmol, seed):
    """BASELINE configs[0]/[1] sized batches (128 molecules = the full config-2 size)."""
    _, edge_graph, _, synth = mods
    b = synth.qm9_batch(nmol, seed=seed)
    ei = torch.from_numpy(b["edge_index"])
    want = ograph.vertex_to_edge_2(ei, len(b["x"]))
    got = edge_graph.vertex_to_edge_2(ei.cuda(), len(b["x"]))
    for a, w in zip(got, want):
        assert torch.equal(a.cpu(), w)
    # size-independent properties: targets non-decreasing; sources ascend inside a segment
    tri = got[0]
    assert bool((tri[1, 1:] >= tri[1, :-1]).all())
    same = tri[1, 1:] == tri[1, :-1]
    assert bool((tri[0, 1:][same] > tri[0, :-1][same]).all())
    # every triplet (f, e) satisfies dst(e) == src(f) and src(e) != dst(f)
    e_src, e_dst = ei[0].cuda(), ei[1].cuda()
    assert bool((e_dst[tri[1]] == e_src[tri[0]]).all()) and bool((e_src[tri[1]] != e_dst[tri[0]]).all())


def test_triplets_random_directed_multigraph_free(mods):
    _, edge_graph, _, _ = mods
    rng = np.random.default_rng(0)
    n = 40
    adj = rng.random((n, n)) < 0.2
    np.fill_diagonal(adj, False)
    ei = np.argwhere(adj).T
    ei = ei[:, rng.permutation(ei.shape[1])]                       # unsorted, directed
    ei = torch.from_numpy(ei.astype(np.int64))
    want = ograph.vertex_to_edge_2_bruteforce(ei, n)
    got = edge_graph.vertex_to_edge_2(ei, n)
    for a, w in zip(got, want):
        assert torch.equal(a.cpu(), w)


def test_radius_graph_batched(mods):
    atom_graph, _, _, synth = mods
    b = synth.qm9_batch(16, seed=3)
    pos = torch.from_numpy(b["atom_pos"]).cuda()
    batch = torch.from_numpy(b["batch"]).cuda()
    ei, edge_num = atom_graph.radius_graph(pos, batch, 5.0)
    assert torch.equal(ei.cpu(), torch.from_numpy(b["edge_index"]))
    assert torch.equal(edge_num.cpu(), torch.from_numpy(b["edge_num"]))
    big = synth.ball_batch(1, n_atoms=500, seed=0)                  # BASELINE configs[3] size
    ei, _ = atom_graph.radius_graph(torch.from_numpy(big["atom_pos"]).cuda(), None, 5.0)
    assert torch.equal(ei.cpu(), torch.from_numpy(big["edge_index"]))
    # symmetric, no self loops, sorted
    assert bool((ei[0] != ei[1]).all())
    key = ei[0] * 500 + ei[1]
    assert bool((key[1:] > key[:-1]).all())
    rev = ei[1] * 500 + ei[0]
    assert torch.equal(torch.sort(rev).values, key)


def test_line_graph_meta(mods):
    _, edge_graph, graph_meta, synth = mods
    b = synth.qm9_batch(6, seed=5)
    ei = torch.from_numpy(b["edge_index"]).cuda()
    E = ei.size(1)
    tri = edge_graph.vertex_to_edge_2(ei, len(b["x"]))[0]
    m = graph_meta.build(tri, E)
    assert m.target_sorted and m.T == tri.size(1)
    T = m.T
    assert torch.equal(m.src.long(), tri[0]) and torch.equal(m.tgt.long(), tri[1])
    cnt_t = torch.bincount(tri[1], minlength=E)
    assert torch.equal(m.rowptr_tgt.long(), torch.cat([cnt_t.new_zeros(1), cnt_t.cumsum(0)]))
    assert torch.equal(m.order_tgt.long(), torch.arange(T, device="cuda"))
    cnt_s = torch.bincount(tri[0], minlength=E)
    assert torch.equal(m.rowptr_src.long(), torch.cat([cnt_s.new_zeros(1), cnt_s.cumsum(0)]))
    # order_src = stable sort of triplet ids by source
    want = torch.sort(tri[0], stable=True).indices
    assert torch.equal(m.order_src.long(), want)
    # unsorted input: order_tgt is the stable sort by target
    perm = torch.randperm(T, device="cuda", generator=torch.Generator("cuda").manual_seed(0))
    tri_u = tri[:, perm].contiguous()
    mu = graph_meta.build(tri_u, E)
    assert not mu.target_sorted
    assert torch.equal(mu.order_tgt.long(), torch.sort(tri_u[1], stable=True).indices)
    assert torch.equal(mu.order_src.long(), torch.sort(tri_u[0], stable=True).indices)
    assert torch.equal(mu.rowptr_tgt, m.rowptr_tgt)
    with pytest.raises(IndexError):
        graph_meta.build(torch.tensor([[0, E], [0, 0]], device="cuda"), E)
    # cache: same tensor object -> same meta object
    assert graph_meta.get(tri, E) is graph_meta.get(tri, E)


def test_device_side_collation_matches_host_collation():
    """x2gnn_b200.collate.DeviceDataset (x2_collate_sizes / x2_collate_fill) against the host collation of the same
    molecules in the same order (PyG semantics: concatenation, per-graph atom offset on edge_index, batch ids,
    edge_num) -- bit-exact; repeated and permuted ids; the collated record runs through the harness model."""
    import numpy as np
    from x2gnn_b200 import synth
    from x2gnn_b200.collate import DeviceDataset
    rng = np.random.default_rng(5)
    mols = []
    for i in range(12):
        pos, z = synth.synth_mol(int(rng.integers(3, 25)), rng)
        ei = synth.radius_edges(pos)
        mols.append({"x": z, "atom_pos": pos.astype(np.float32), "edge_index": ei,
                     "edge_attr": rng.normal(size=(ei.shape[1], 338)).astype(np.float32) * 0.1, "y": float(i)})
    ds = DeviceDataset.from_molecules(mols, "cuda")
    for ids in ([3, 1, 7], [0], [11, 11, 2, 5, 5], list(range(12))):
        got = ds.collate(ids)
        off, xs, ps, eis, fs, bt, en = 0, [], [], [], [], [], []
        for g, m in enumerate(mols[i] for i in ids):
            xs.append(m["x"]); ps.append(m["atom_pos"]); fs.append(m["edge_attr"]); eis.append(m["edge_index"] + off)
            bt.append(np.full(len(m["x"]), g)); en.append(m["edge_index"].shape[1]); off += len(m["x"])
        assert torch.equal(got["x"].cpu(), torch.from_numpy(np.concatenate(xs)).long())
        assert torch.equal(got["atom_pos"].cpu(), torch.from_numpy(np.concatenate(ps)))
        assert torch.equal(got["edge_attr"].cpu(), torch.from_numpy(np.concatenate(fs)))
        assert torch.equal(got["edge_index"].cpu(), torch.from_numpy(np.concatenate(eis, axis=1)).long())
        assert torch.equal(got["batch"].cpu(), torch.from_numpy(np.concatenate(bt)).long())
        assert got["edge_num"].cpu().tolist() == en and got["num_graphs"] == len(ids)
        assert got["y"].cpu().tolist() == [float(i) for i in ids]
    with pytest.raises(IndexError):
        ds.collate([0, 12])
    # the stored bonds are the radius graph of the collated atoms (lexicographic, as gen_bonds_mini produces them)
    from x2gnn_b200 import atom_graph
    got = ds.collate([4, 9, 2])
    ei2, en2 = atom_graph.radius_graph(got["atom_pos"], got["batch"], 5.0)
    assert torch.equal(ei2, got["edge_index"]) and torch.equal(en2, got["edge_num"])
    from x2gnn_b200.xgnn_model import XGNNPoly
    net = XGNNPoly(conv_layers=1, sbf_dim=7, rbf_dim=6, in_channels=128, heads=16, embedding_size=128).cuda()
    assert torch.isfinite(net(got)).all()
